"""Command line of the drop-in: the reference invocation

    python Whisper/Inference_Whisper_ONNX.py --onnx-folder DIR [--tokenizer-path P]

(/root/reference/Whisper/Inference_Whisper_ONNX.py:37-53) becomes

    python -m b200asr.cli whisper --model-folder DIR [--tokenizer-path P] --audio clip.wav [more.wav ...]
    python -m b200asr.cli qwen    --model-folder DIR [--tokenizer-path P] --audio clip.wav [--language English] [--prompt "..."]
    python -m b200asr.cli sensevoice --model-folder DIR [--tokenizer-path model.bpe] --audio clip.wav [--language auto]
    python -m b200asr.cli paraformer --model-folder DIR [--vocab-path tokens.json] --audio clip.wav

(the second = /root/reference/Qwen_ASR/Inference_Qwen_ASR_ONNX.py:44-60; DIR = the Qwen3-ASR checkpoint folder with the
tokenizer files, REPEAT_PENALTY / PENALTY_RANGE via --set as in the script's configuration block :84-91)

where DIR is the HF checkpoint folder the exporter starts from (`config.json`, `model.safetensors`,
`generation_config.json`; Export_Whisper.py:14).  Behaviour constants keep the script's names (`--set REPEAT_PENALTY=1.0
DETECT_LANGUAGE=0 ...`, defaults of :71-100).  Output = the script's block: `ASR Result:` / `RTF:` (:836-841).
Without tokenizer files the token ids are printed instead of text.
"""
from __future__ import annotations

import argparse
import json
import sys
from dataclasses import fields
from pathlib import Path

import numpy as np

from . import ingest
from .engine import WhisperEngine
from .weights import fold_whisper
from .whisper_infer import InferenceOptions, WhisperPipeline


def whisper_metadata(dims, gen: dict, sample_rate: int = 16000) -> dict:
    """The custom_metadata_map the exporter writes into ASR_Metadata.onnx (Export_Whisper.py:684-716, 1062-1074), from the
    checkpoint's generation config."""
    lang = gen.get("lang_to_id") or {}
    langs = {tok[2:-2]: {"name": tok[2:-2], "aliases": [], "token_id": int(i), "prompt_token_ids": []} for tok, i in lang.items()}
    no_ts = int(gen.get("no_timestamps_token_id", 0))
    special = {"decoder_start": int(gen.get("decoder_start_token_id", 0)), "eos": int(gen.get("eos_token_id", 0)),
               "stop": [int(gen.get("eos_token_id", 0))], "no_speech": int(gen.get("no_speech_token_id", no_ts - 1)),
               "no_timestamps": no_ts, "tasks": {str(k): int(v) for k, v in (gen.get("task_to_id") or {}).items()}}
    return {"audio_pcm_scale": "32768", "max_seq_len": str(dims.max_target), "sample_rate": str(sample_rate),
            "special_token_ids": json.dumps(special), "supported_languages": json.dumps(langs)}


def _options(pairs) -> InferenceOptions:
    opt = InferenceOptions()
    types = {f.name: f.type for f in fields(InferenceOptions)}
    for p in pairs or []:
        k, _, v = p.partition("=")
        if k not in types:
            raise SystemExit(f"unknown option {k}; choose from {sorted(types)}")
        cur = getattr(opt, k)
        setattr(opt, k, (v.lower() in ("1", "true", "yes")) if isinstance(cur, bool) else type(cur)(v))
    return opt


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="b200asr")
    sub = ap.add_subparsers(dest="model", required=True)
    w = sub.add_parser("whisper")
    w.add_argument("--model-folder", "--onnx-folder", dest="folder", required=True)
    w.add_argument("--tokenizer-path", default=None)
    w.add_argument("--audio", nargs="+", required=True)
    w.add_argument("--precision", default="bf16", choices=["bf16", "f32"])
    w.add_argument("--device", type=int, default=0)
    w.add_argument("--set", nargs="*", default=[], metavar="NAME=VALUE", help="script constants, e.g. REPEAT_PENALTY=1.0")
    w.add_argument("--batch", type=int, default=1, help="transcribe the files in ragged batches of up to N clips (single-window clips; "
                                                        "every clip keeps its single-clip result)")
    qp = sub.add_parser("qwen")
    qp.add_argument("--model-folder", "--onnx-folder", dest="folder", required=True)
    qp.add_argument("--tokenizer-path", default=None)
    qp.add_argument("--audio", nargs="+", required=True)
    qp.add_argument("--language", default="", help="force a language (its name + <asr_text> is appended to the prompt); empty = detect")
    qp.add_argument("--prompt", default="", help="task / context prompt placed in the system turn")
    qp.add_argument("--precision", default="bf16", choices=["bf16", "f32"])
    qp.add_argument("--device", type=int, default=0)
    qp.add_argument("--set", nargs="*", default=[], metavar="NAME=VALUE", help="REPEAT_PENALTY=1.0 PENALTY_RANGE=10")
    qp.add_argument("--batch", type=int, default=1, help="transcribe the files in ragged batches of up to N clips (each clip keeps its single-clip result)")
    svp = sub.add_parser("sensevoice")
    svp.add_argument("--model-folder", "--onnx-folder", dest="folder", required=True)
    svp.add_argument("--tokenizer-path", default=None, help="SentencePiece model (default: chn_jpn_yue_eng_ko_spectok.bpe.model in the folder)")
    svp.add_argument("--audio", nargs="+", required=True)
    svp.add_argument("--language", default="auto")
    svp.add_argument("--precision", default="bf16", choices=["bf16", "f32"])
    svp.add_argument("--device", type=int, default=0)
    svp.add_argument("--set", nargs="*", default=[], metavar="NAME=VALUE", help="INPUT_AUDIO_LENGTH=0 SLIDING_WINDOW=0 (0 = dynamic axis / window stride)")
    pfp = sub.add_parser("paraformer")
    pfp.add_argument("--model-folder", "--onnx-folder", dest="folder", required=True)
    pfp.add_argument("--vocab-path", "--tokenizer-path", dest="tokenizer_path", default=None, help="tokens.json / Vocab_Paraformer.txt (default: looked up in the folder)")
    pfp.add_argument("--audio", nargs="+", required=True)
    pfp.add_argument("--precision", default="bf16", choices=["bf16", "f32"])
    pfp.add_argument("--device", type=int, default=0)
    pfp.add_argument("--set", nargs="*", default=[], metavar="NAME=VALUE", help="INPUT_AUDIO_LENGTH=0 SLIDING_WINDOW=0 DECODE_MODE=zh")
    args = ap.parse_args(argv)
    if args.model == "qwen":
        return _main_qwen(args)
    if args.model == "sensevoice":
        return _main_sensevoice(args)
    if args.model == "paraformer":
        return _main_paraformer(args)

    dims, state, gen = ingest.load_hf_whisper(args.folder)
    tensors = fold_whisper(state, dims, gen.get("suppress_tokens") or [], gen.get("begin_suppress_tokens") or [])
    del state
    md = whisper_metadata(dims, gen)
    md_path = Path(args.folder) / "ASR_Metadata.onnx"
    if md_path.exists():                     # an exported folder's own run-time constants win (Inference_Whisper_ONNX.py:270-289)
        from . import onnx_io
        md.update(onnx_io.read_metadata(md_path))
    tokenizer = None
    tok_dir = Path(args.tokenizer_path) if args.tokenizer_path else Path(args.folder)
    try:
        from transformers import AutoTokenizer
        tokenizer = AutoTokenizer.from_pretrained(str(tok_dir))
    except Exception as exc:                                   # no tokenizer files: ids are still a complete result
        print(f"(tokenizer not loaded from {tok_dir}: {exc.__class__.__name__}; printing token ids)", file=sys.stderr)
    clips = [ingest.read_wav(p) for p in args.audio]
    sr = int(md["sample_rate"])
    pcm = [ingest.to_model_rate(x, r, sr) for x, r in clips]
    opts = _options(args.set)
    nb = max(1, min(args.batch, 8, len(pcm)))
    if opts.INPUT_AUDIO_LENGTH > 0 and max(len(x) for x in pcm) > opts.INPUT_AUDIO_LENGTH:
        nb = 8                               # long-form: the windows after the first run as batches (WhisperPipeline, BATCH_WINDOWS)
    eng = WhisperEngine(dims, tensors, precision=args.precision, max_batch=nb, max_samples=max(480000, max(len(x) for x in pcm)),
                        device=args.device)
    pipe = WhisperPipeline(eng, md, opts)
    if nb > 1 and opts.INPUT_AUDIO_LENGTH <= 0:
        # ragged batches: neighbours in length share a batch (the batch runs at its longest clip's row count)
        from .sharding import ragged_batches
        results = {}
        for group in ragged_batches(range(len(pcm)), [len(x) for x in pcm], nb):
            for i, res in zip(group, pipe.transcribe_batch([pcm[i] for i in group])):
                results[i] = res
        for i, path in enumerate(args.audio):
            res = results[i]
            print("-" * 106)
            print(f"\nTest Input Audio: {path}")
            print(f"Detected Language: {res.language}")
            text = tokenizer.decode(res.tokens, skip_special_tokens=True) if tokenizer is not None else " ".join(map(str, res.tokens))
            print(pipe.report(res, text))
        eng.close()
        return 0
    for path, x in zip(args.audio, pcm):
        print("-" * 106)
        print(f"\nTest Input Audio: {path}")
        res = pipe.transcribe_pcm(x, verbose=True)
        text = tokenizer.decode(res.tokens, skip_special_tokens=True) if tokenizer is not None else " ".join(map(str, res.tokens))
        print(pipe.report(res, text))
    eng.close()
    return 0


def _main_qwen(args) -> int:
    """Inference_Qwen_ASR_ONNX.py main(): checkpoint folder + tokenizer -> per clip `ASR Result` / `RTF` (:745-760)."""
    from . import qwen as qw
    from transformers import AutoTokenizer
    consts = {"REPEAT_PENALTY": qw.REPEAT_PENALTY, "PENALTY_RANGE": qw.PENALTY_RANGE}
    for pair in args.set or []:
        k, _, v = pair.partition("=")
        if k not in consts:
            raise SystemExit(f"unknown option {k}; choose from {sorted(consts)}")
        consts[k] = type(consts[k])(v)
    dims, state, tied = ingest.load_hf_qwen3_asr(args.folder)
    tokenizer = AutoTokenizer.from_pretrained(str(args.tokenizer_path or args.folder))      # prompt ids come from the tokenizer: required
    prompt, tails = ingest.qwen_prompt_from_tokenizer(tokenizer, [args.language] if args.language else [])
    tensors = qw.fold_qwen(state, dims, tie_lm_head=tied)
    del state
    clips = [ingest.read_wav(p) for p in args.audio]
    pcm = [ingest.to_model_rate(x, r, dims.sample_rate) for x, r in clips]
    nb = max(1, min(args.batch, 8, len(pcm)))
    eng = qw.QwenEngine(dims, tensors, prompt, precision=args.precision, max_batch=nb,
                        max_samples=max(480000, max(len(x) for x in pcm)), device=args.device)
    query = [int(i) for i in tokenizer.encode(args.prompt, add_special_tokens=False)] if args.prompt else []
    kw = dict(query_ids=query, language_tail_ids=tails.get(args.language, ()), sample_rate=dims.sample_rate,
              repeat_penalty=float(consts["REPEAT_PENALTY"]), penalty_range=int(consts["PENALTY_RANGE"]))
    batched = qw.transcribe_clips(eng, pcm, max_batch=nb, **kw) if nb > 1 else None
    for k, (path, x) in enumerate(zip(args.audio, pcm)):
        print(f"\nTest audio : {path}   ({len(x) / dims.sample_rate:.2f} s)")
        print("-" * 70)
        res = batched[k] if batched is not None else qw.transcribe_clip(eng, x, **kw)
        text = tokenizer.decode(res["tokens"], skip_special_tokens=True)
        print(f"\nASR Result:\n{text}\n\nRTF: {res['rtf']:.4f}")
    eng.close()
    return 0


def _main_sensevoice(args) -> int:
    """Inference_SenseVoice_ONNX.py: FunASR folder (`model.pt`, `am.mvn`, SentencePiece model) -> per clip `ASR Result` / `RTF`
    (:236-310); INPUT_AUDIO_LENGTH / SLIDING_WINDOW as in the script's configuration block."""
    from . import sensevoice as sv
    consts = {"INPUT_AUDIO_LENGTH": 0, "SLIDING_WINDOW": 0}
    for pair in args.set or []:
        k, _, v = pair.partition("=")
        if k not in consts:
            raise SystemExit(f"unknown option {k}; choose from {sorted(consts)}")
        consts[k] = int(v)
    dims, raw = ingest.load_funasr_sensevoice(args.folder)
    clips = [ingest.read_wav(p) for p in args.audio]
    pcm = [ingest.to_model_rate(x, r, dims.sample_rate) for x, r in clips]
    win = consts["INPUT_AUDIO_LENGTH"] or None
    max_samples = win or max(len(x) for x in pcm)
    eng = sv.SenseVoiceEngine(dims, sv.fold_sensevoice(raw, dims, max_samples), precision=args.precision, max_batch=8,
                              max_samples=max_samples, device=args.device)
    sp = None
    tok_path = Path(args.tokenizer_path) if args.tokenizer_path else Path(args.folder) / "chn_jpn_yue_eng_ko_spectok.bpe.model"
    try:
        import sentencepiece
        sp = sentencepiece.SentencePieceProcessor(model_file=str(tok_path))
    except Exception as exc:
        print(f"(tokenizer not loaded from {tok_path}: {exc.__class__.__name__}; printing token ids)", file=sys.stderr)
    for path, x in zip(args.audio, pcm):
        print("-" * 106)
        print(f"\nTest Input Audio: {path}")
        res = sv.transcribe_long(eng, x, args.language, input_audio_length=win, sliding_window=consts["SLIDING_WINDOW"],
                                 sample_rate=dims.sample_rate)
        # the script detokenises window by window and concatenates the text (Inference_SenseVoice_ONNX.py:303-305); SentencePiece
        # strips the leading word-boundary mark of every decode call, so decoding the concatenated ids would add a space per window
        text = ("".join(sp.decode(w) for w in res["per_window"]) if sp is not None else " ".join(map(str, res["tokens"])))
        print(f"\nASR Result:\n{text}\n\nRTF: {res['rtf']:.4f}\n")
    eng.close()
    return 0


def _main_paraformer(args) -> int:
    """Paraformer/Non-Streaming/Inference_Paraformer_ONNX.py: FunASR folder + vocabulary -> per clip `ASR Result` / `RTF` (:233-302)."""
    from . import paraformer as pfm
    from . import sensevoice as sv
    consts = {"INPUT_AUDIO_LENGTH": 0, "SLIDING_WINDOW": 0, "DECODE_MODE": "zh"}
    for pair in args.set or []:
        k, _, v = pair.partition("=")
        if k not in consts:
            raise SystemExit(f"unknown option {k}; choose from {sorted(consts)}")
        consts[k] = type(consts[k])(v)
    dims, raw = ingest.load_funasr_paraformer(args.folder)
    clips = [ingest.read_wav(p) for p in args.audio]
    pcm = [ingest.to_model_rate(x, r, dims.sample_rate) for x, r in clips]
    win = consts["INPUT_AUDIO_LENGTH"] or None
    max_samples = win or max(len(x) for x in pcm)
    eng = pfm.ParaformerEngine(dims, pfm.fold_paraformer(raw, dims, max_samples), precision=args.precision, max_batch=8,
                               max_samples=max_samples, device=args.device)
    vocab = None
    cands = [Path(args.tokenizer_path)] if args.tokenizer_path else [Path(args.folder) / n for n in ("Vocab_Paraformer.txt", "tokens.json", "tokens.txt")]
    for cand in cands:
        if cand.exists():
            vocab = ingest.read_vocab(cand)
            break
    if vocab is None:
        print("(vocabulary not found; printing token ids)", file=sys.stderr)
    for path, x in zip(args.audio, pcm):
        print("-" * 106)
        print(f"\nTest Input Audio: {path}")
        res = sv.transcribe_long(eng, x, input_audio_length=win, sliding_window=consts["SLIDING_WINDOW"], sample_rate=dims.sample_rate)
        # per window, then concatenated, as the script does
        text = ("".join(pfm.tokens_to_text(w, vocab, consts["DECODE_MODE"]) for w in res["per_window"]).strip()
                if vocab is not None else " ".join(map(str, res["tokens"])))
        print(f"\nASR Result:\n{text}\n\nRTF: {res['rtf']:.4f}\n")
    eng.close()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())

"""Load the reference's own graph-defining modules (DEV CONTAINER ONLY).

TEST INFRASTRUCTURE.  /root/reference does not exist on the GPU box; nothing
that runs there imports this file.  It exists so ``gen_golden.py`` can mint
golden vectors from the *reference implementation itself*:
``Whisper/Export_Whisper.py`` cannot be imported (module-level code loads a
checkpoint from ~/Downloads and exports ONNX), so the needed ClassDef /
FunctionDef nodes are AST-extracted and exec'd into a namespace seeded with the
module constants, exactly as SURVEY.md section 8c describes.  No reference
source is copied into this repository.
"""
from __future__ import annotations

import ast
import sys
from pathlib import Path

import torch
import torchaudio

REF_ROOT = Path("/root/reference")

_WANT = {
    "_bias_or_zero", "absorb_layer_norm_affine", "WHISPER_ENCODER", "WHISPER_DECODER",
    "WHISPER_DECODER_EMBED", "WHISPER_PREFILL", "WHISPER_DECODE", "BEGIN_SUPPRESS", "ARGMAX",
    "GREEDY_SEARCH", "APPLY_PENALTY", "NO_SPEECH_DETECTION",
}


def reference_available() -> bool:
    return (REF_ROOT / "Whisper" / "Export_Whisper.py").exists()


STAGED = Path(__file__).resolve().parent / "_ref" / "whisper_ref.bin"      # written by oracle/stage_ref.py (git-ignored)


def staged_available() -> bool:
    return STAGED.exists()


def patched_export_source(src: str) -> str:
    # Export_Whisper.py:619 only works under tracing (shape[0] is an int in eager mode).
    return src.replace("batch_size = hidden_states.shape[0].unsqueeze(0)",
                       "batch_size = hidden_states.shape[0]")


def _seed_namespace(use_fp16_kv: bool):
    # the module-level constants the extracted definitions read (Export_Whisper.py configuration block)
    return dict(torch=torch, torchaudio=torchaudio, INPUT_AUDIO_DTYPE="F32",
                USE_FP16_KV=use_fp16_kv, COMPUTE_IN_F32=False,
                KV_DTYPE=torch.float16 if use_fp16_kv else torch.float32,
                REORDER_DOWNPROJ_FOR_QUANT=False, REORDER_OPROJ_FOR_QUANT=False, REORDER_KEY="absmean")


def load_whisper_namespace(use_fp16_kv: bool = False):
    src = patched_export_source((REF_ROOT / "Whisper" / "Export_Whisper.py").read_text())
    body = [n for n in ast.parse(src).body
            if isinstance(n, (ast.ClassDef, ast.FunctionDef)) and n.name in _WANT]
    ns = _seed_namespace(use_fp16_kv)
    exec(compile(ast.Module(body=body, type_ignores=[]), "ref_whisper", "exec"), ns)
    sys.path.insert(0, str(REF_ROOT / "Whisper"))
    try:
        from STFT_Process import STFT_Process  # type: ignore
    finally:
        sys.path.pop(0)
    ns["STFT_Process"] = STFT_Process
    return ns


def load_staged_namespace(use_fp16_kv: bool = False):
    """The same namespace from oracle/_ref/whisper_ref.bin (the reference's definitions compiled by oracle/stage_ref.py):
    works on the GPU box, where /root/reference does not exist."""
    import json
    import marshal
    meta, code_defs, code_stft = marshal.loads(STAGED.read_bytes())
    meta = json.loads(meta)
    if tuple(meta["python"][:2]) != tuple(sys.version_info[:2]):
        raise RuntimeError(f"oracle/_ref was staged under Python {meta['python']}, this is {sys.version_info[:3]}")
    ns = _seed_namespace(use_fp16_kv)
    exec(code_defs, ns)
    stft_ns = {"__name__": "STFT_Process"}
    exec(code_stft, stft_ns)
    ns["STFT_Process"] = stft_ns["STFT_Process"]
    ns["__staged_meta__"] = meta
    return ns


def build_reference_whisper(raw_weights, dims, suppress_tokens, begin_suppress_tokens, staged: bool = False):
    """Instantiate HF Whisper with ``raw_weights`` and wrap it in the reference modules."""
    from transformers import WhisperConfig, WhisperForConditionalGeneration

    ns = load_staged_namespace() if staged else load_whisper_namespace()
    cfg = WhisperConfig(
        vocab_size=dims.vocab, num_mel_bins=dims.n_mels, d_model=dims.d_model,
        encoder_layers=dims.enc_layers, decoder_layers=dims.dec_layers,
        encoder_attention_heads=dims.n_heads, decoder_attention_heads=dims.n_heads,
        encoder_ffn_dim=dims.ffn, decoder_ffn_dim=dims.ffn,
        max_source_positions=dims.max_source, max_target_positions=dims.max_target,
        pad_token_id=0, bos_token_id=1, eos_token_id=2, decoder_start_token_id=3)
    model = WhisperForConditionalGeneration(cfg).eval()
    missing, unexpected = model.load_state_dict(raw_weights, strict=False)
    assert not unexpected, unexpected
    assert all("k_proj.bias" in m for m in missing), missing
    stft = ns["STFT_Process"]('stft_B_power', dims.n_fft, dims.n_fft, dims.hop, 0, 'hann',
                              center_pad=True, pad_mode='reflect', input_scale=1.0,
                              drop_last_frame=True).eval()
    with torch.no_grad():
        enc = ns["WHISPER_ENCODER"](model.model, stft, dims.n_fft, dims.n_mels, dims.sample_rate,
                                    dims.dec_layers).eval()          # FIRST (deletes cross k/v_proj)
        sup = None if suppress_tokens is None else torch.tensor(list(suppress_tokens), dtype=torch.int64)
        dec = ns["WHISPER_DECODER"](model, sup, dims.dec_layers).eval()
        mods = dict(
            ns=ns, model=model, stft=stft, encoder=enc, decoder=dec,
            embed=ns["WHISPER_DECODER_EMBED"](model.model.decoder).eval(),
            prefill=ns["WHISPER_PREFILL"](model.model.decoder, dims.max_target, torch.float32).eval(),
            decode=ns["WHISPER_DECODE"](model.model.decoder).eval(),
            begin=ns["BEGIN_SUPPRESS"](tuple(begin_suppress_tokens), dims.vocab).eval(),
            argmax=ns["ARGMAX"]().eval(),
            greedy=ns["GREEDY_SEARCH"]().eval(),
            penalty=ns["APPLY_PENALTY"]().eval(),
            no_speech=None,
        )
    return mods


def reference_greedy(mods, dims, pcm_int16, prompt, max_new):
    """The reference modules driven the way Whisper/Inference_Whisper_ONNX.py:437-663 drives its sessions (greedy, no stop
    token): 1 encoder launch, 1 prefill launch over the prompt, max_new - 1 decode launches with the self-KV fed back.
    Returns the selected ids (begin-suppress on the first head only)."""
    import numpy as np
    L = dims.dec_layers
    enc, dec = mods["encoder"], mods["decoder"]
    a = np.asarray(pcm_int16).astype(np.float32) * np.float32(1.0 / 32768.0)       # prepare_audio_input, F32 input
    audio = torch.from_numpy(a).reshape(1, 1, -1)
    with torch.no_grad():
        cross = enc(audio)
        ck, cv = list(cross[:L]), list(cross[L:])
        sk = [torch.zeros(1, dims.n_heads, dims.head_dim, 0) for _ in range(L)]
        sv = [torch.zeros(1, dims.n_heads, 0, dims.head_dim) for _ in range(L)]
        n = len(prompt)
        emb = mods["embed"](torch.tensor([list(prompt)], dtype=torch.int32))
        pe, mask, _ = mods["prefill"](torch.tensor([n]), torch.tensor([0]))
        r = dec(*sk, *sv, *ck, *cv, emb, pe, mask)
        sk, sv, logits = list(r[:L]), list(r[L:2 * L]), r[-1]
        tok = int(mods["argmax"](mods["begin"](logits))[0, 0])
        toks = [tok]
        hist = n
        zero_mask = torch.zeros(1, 1, 1)
        while len(toks) < max_new:
            emb = mods["embed"](torch.tensor([[tok]], dtype=torch.int32))
            pe, _ = mods["decode"](torch.tensor([hist]))
            r = dec(*sk, *sv, *ck, *cv, emb, pe, zero_mask)
            sk, sv, logits = list(r[:L]), list(r[L:2 * L]), r[-1]
            hist += 1
            tok = int(mods["argmax"](logits)[0, 0])
            toks.append(tok)
    return toks

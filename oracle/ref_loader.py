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


def load_whisper_namespace(use_fp16_kv: bool = False):
    src = (REF_ROOT / "Whisper" / "Export_Whisper.py").read_text()
    # Export_Whisper.py:619 only works under tracing (shape[0] is an int in eager mode).
    src = src.replace("batch_size = hidden_states.shape[0].unsqueeze(0)",
                      "batch_size = hidden_states.shape[0]")
    body = [n for n in ast.parse(src).body
            if isinstance(n, (ast.ClassDef, ast.FunctionDef)) and n.name in _WANT]
    ns = dict(torch=torch, torchaudio=torchaudio, INPUT_AUDIO_DTYPE="F32",
              USE_FP16_KV=use_fp16_kv, COMPUTE_IN_F32=False,
              KV_DTYPE=torch.float16 if use_fp16_kv else torch.float32,
              REORDER_DOWNPROJ_FOR_QUANT=False, REORDER_OPROJ_FOR_QUANT=False, REORDER_KEY="absmean")
    exec(compile(ast.Module(body=body, type_ignores=[]), "ref_whisper", "exec"), ns)
    sys.path.insert(0, str(REF_ROOT / "Whisper"))
    try:
        from STFT_Process import STFT_Process  # type: ignore
    finally:
        sys.path.pop(0)
    ns["STFT_Process"] = STFT_Process
    return ns


def build_reference_whisper(raw_weights, dims, suppress_tokens, begin_suppress_tokens):
    """Instantiate HF Whisper with ``raw_weights`` and wrap it in the reference modules."""
    from transformers import WhisperConfig, WhisperForConditionalGeneration

    ns = load_whisper_namespace()
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

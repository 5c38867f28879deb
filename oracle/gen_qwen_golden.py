"""Mint Qwen3-ASR golden vectors from the REFERENCE classes themselves (DEV CONTAINER ONLY).

/root/reference/Qwen_ASR/Export_Qwen_ASR.py cannot be imported (module-level code loads a checkpoint and exports), so
its ClassDef / FunctionDef nodes are AST-extracted and exec'd with the module constants (fp32 KV, fp32 rotary storage,
quantisation reorders off -- they are exact permutations absorbed into the weights).  The skeleton model (:311-516) is
instantiated from a tiny config, filled with the oracle's seeded checkpoint, and run through QWEN3_ASR_ENCODER,
QWEN3_ASR_ROTARY_MASK_PREFILL/_DECODE, QWEN3_ASR_DECODER_MAIN and ARGMAX exactly as Inference_Qwen_ASR_ONNX.py:656-737
chains them.  Before anything is written the CPU oracle (oracle/qwen_oracle.py) must reproduce every stage.
No reference source is copied.  Outputs -> tests/golden/qwen_tiny_case*.npz
"""
import ast
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F
import torchaudio

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import qwen_oracle as qo  # noqa: E402

REF_DIR = Path("/root/reference/Qwen_ASR")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def load_namespace(max_audio_len: int):
    src = (REF_DIR / "Export_Qwen_ASR.py").read_text()
    tree = ast.parse(src)
    skip = {"build_model_metadata", "replace_onnx_metadata", "refresh_non_persistent_buffers", "get_kv_io"}
    body = [n for n in tree.body if isinstance(n, (ast.ClassDef, ast.FunctionDef)) and n.name not in skip]
    from torch import Tensor, nn
    from torch.onnx import symbolic_helper
    from transformers import AutoConfig, AutoModel
    from transformers.activations import ACT2FN
    from transformers.configuration_utils import PretrainedConfig
    from transformers.generation import GenerationMixin
    from transformers.modeling_layers import GradientCheckpointingLayer
    from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS
    from transformers.modeling_utils import PreTrainedModel
    from typing import Dict, List, Sequence, Tuple
    sys.path.insert(0, str(REF_DIR))
    try:
        from STFT_Process import STFT_Process  # type: ignore
    finally:
        sys.path.pop(0)
    ns = dict(torch=torch, np=np, F=F, torchaudio=torchaudio, Tensor=Tensor, nn=nn, symbolic_helper=symbolic_helper,
              AutoConfig=AutoConfig, AutoModel=AutoModel, ACT2FN=ACT2FN, PretrainedConfig=PretrainedConfig,
              GenerationMixin=GenerationMixin, GradientCheckpointingLayer=GradientCheckpointingLayer,
              ROPE_INIT_FUNCTIONS=ROPE_INIT_FUNCTIONS, PreTrainedModel=PreTrainedModel, Dict=Dict, List=List,
              Sequence=Sequence, Tuple=Tuple, STFT_Process=STFT_Process,
              MAX_INPUT_AUDIO_LENGTH=max_audio_len, USE_FP16_KV=False, COMPUTE_IN_F32=False,
              ROTARY_STORAGE_DTYPE=torch.float32, INPUT_AUDIO_DTYPE="F32", REORDER_DOWNPROJ_FOR_QUANT=False,
              REORDER_OPROJ_FOR_QUANT=False, REORDER_KEY="absmean", _MODEL_SAMPLE_RATE=16000, _MODEL_WINDOW_TYPE="hann",
              _MODEL_NUM_MELS=128, _MODEL_NFFT_STFT=400, _MODEL_WINDOW_LENGTH=400, _MODEL_HOP_LENGTH=160,
              _MODEL_AUDIO_PCM_SCALE=32768)
    exec(compile(ast.Module(body=body, type_ignores=[]), "ref_qwen_asr", "exec"), ns)
    return ns


def build_reference_model(ns, d: qo.QwenDims, raw):
    cfg = ns["Qwen3ASRConfig"](thinker_config=dict(
        audio_config=dict(num_mel_bins=d.n_mels, encoder_layers=d.enc_layers, encoder_attention_heads=d.enc_heads,
                          encoder_ffn_dim=d.enc_ffn, d_model=d.enc_d, max_source_positions=d.max_source_positions,
                          n_window=qo.CHUNK // 2, output_dim=d.out_dim, n_window_infer=qo.CHUNK * d.chunks_per_window,
                          downsample_hidden_size=d.conv_ch),
        text_config=dict(vocab_size=d.vocab, hidden_size=d.hidden, intermediate_size=d.inter, num_hidden_layers=d.dec_layers,
                         num_attention_heads=d.heads, num_key_value_heads=d.kv_heads, head_dim=d.head_dim,
                         rope_theta=d.rope_theta, rms_norm_eps=d.rms_eps, max_position_embeddings=4096)))
    model = ns["Qwen3ASRForConditionalGeneration"](cfg).eval().float()
    sd = model.state_dict()
    missing = [k for k in sd if k not in raw]
    extra = [k for k in raw if k not in sd]
    assert not missing and not extra, (missing[:5], extra[:5])
    model.load_state_dict({k: v.clone() for k, v in raw.items()})
    # non-persistent buffers the wrappers read (what refresh_non_persistent_buffers :551-583 restores after from_pretrained)
    at = model.thinker.audio_tower
    at.positional_embedding.positional_embedding = qo.sinusoid_positions(d.max_source_positions, d.enc_d)
    dim = d.head_dim
    model.thinker.model.rotary_emb.inv_freq = 1.0 / (d.rope_theta ** (torch.arange(0, dim, 2, dtype=torch.int64).float() / dim))
    return model


def synth_pcm(seed, n):
    g = torch.Generator().manual_seed(1234 + seed)
    x = torch.randn(n, generator=g) * 1638.0
    t = torch.arange(n, dtype=torch.float32) / 16000.0
    for f0 in (220.0, 440.0, 1760.0):
        x = x + 3000.0 * torch.sin(2 * torch.pi * f0 * t * (1.0 + 0.1 * seed))
    return x.round().clamp(-32768, 32767).to(torch.int16).numpy()


def main():
    d = qo.TINY_TEST
    pr_ = qo.TINY_PROMPT
    max_audio = 480000
    ns = load_namespace(max_audio)
    cases = [(0, 32000, (), (), 6), (1, 171360, (20, 21, 22), (30, 31), 5), (2, 130000, (), (40,), 6), (3, 15999, (7,), (), 4)]
    for case, (seed, n, query_ids, lang_tail, n_forced) in enumerate(cases):
        raw = qo.make_raw_weights(d, seed)
        model = build_reference_model(ns, d, raw)
        embed = model.thinker.model.embed_tokens
        with torch.no_grad():
            enc = ns["QWEN3_ASR_ENCODER"](model.thinker.audio_tower, embed.float(), list(pr_.head_ids), list(pr_.tail_ids),
                                          list(pr_.suffix_ids)).eval()
            rot_p = ns["QWEN3_ASR_ROTARY_MASK_PREFILL"](model.thinker.model, d.max_seq_len).eval()
            rot_d = ns["QWEN3_ASR_ROTARY_MASK_DECODE"](model.thinker.model, d.max_seq_len).eval()
            dec = ns["QWEN3_ASR_DECODER_MAIN"](model, d.heads, d.kv_heads, d.head_dim, d.dec_layers, d.hidden).eval()
            concat = ns["CONCAT_EMBED"]().eval()
            argmax = ns["ARGMAX"]().eval()
            pcm = synth_pcm(seed, n)
            audio = qo.prepare_audio(pcm)
            q_emb = embed(torch.tensor([list(query_ids)], dtype=torch.int32)).float() if query_ids else torch.zeros(1, 0, d.hidden)
            t_emb = embed(torch.tensor([list(lang_tail)], dtype=torch.int32)).float() if lang_tail else torch.zeros(1, 0, d.hidden)
            base, _ = enc(audio, q_emb)
            prompt_embed, ids_len = concat(base, t_emb)
            n_prompt = int(ids_len)
            cos, sin, mask, kv_len = rot_p(ids_len, torch.tensor([0], dtype=torch.int64))
            L = d.dec_layers
            kv = [torch.zeros(1, d.kv_heads, 1, d.head_dim, 0) for _ in range(L)] + \
                 [torch.zeros(1, d.kv_heads, 1, 0, d.head_dim) for _ in range(L)]
            out = dec(*kv, prompt_embed, cos, sin, mask)
            logits = [out[-1][0]]
            # free-running greedy stream (Inference_Qwen_ASR_ONNX.py:683-737) for max_new tokens
            max_new = 8
            stop = set(pr_.stop_ids)
            tok = int(argmax(out[-1]))
            tokens, count = [], 0
            state, kvl = list(out[:2 * L]), kv_len
            if tok not in stop:
                count = 1; tokens.append(tok)
            while count < max_new and tok not in stop:
                c1, s1, kvl = rot_d(kvl)
                o = dec(*state, embed(torch.tensor([[tok]], dtype=torch.int32)).float(), c1, s1, torch.zeros(1, 1, 1, 1, 1))
                state = list(o[:2 * L])
                tok = int(argmax(o[-1]))
                if tok not in stop:
                    count += 1; tokens.append(tok)
            # the script's default strategy: penalty-greedy (REPEAT_PENALTY 0.8, PENALTY_RANGE 10; Inference_Qwen_ASR_ONNX.py:90-91)
            greedy_head, pen_head = ns["GREEDY_SEARCH"]().eval(), ns["APPLY_PENALTY"]().eval()
            pv, pr = torch.tensor(0.8), torch.tensor(10, dtype=torch.int64)
            p_new = 14
            tok_t, save = greedy_head(out[-1], torch.zeros((1, 0), dtype=torch.int32))
            tok = int(tok_t)
            pen_tokens, count = [], 0
            state, kvl = list(out[:2 * L]), kv_len
            if tok not in stop:
                count = 1; pen_tokens.append(tok)
            while count < p_new and tok not in stop:
                c1, s1, kvl = rot_d(kvl)
                o = dec(*state, embed(torch.tensor([[tok]], dtype=torch.int32)).float(), c1, s1, torch.zeros(1, 1, 1, 1, 1))
                state = list(o[:2 * L])
                tok_t, save = greedy_head(pen_head(o[-1], save, pv, pr), save)
                tok = int(tok_t)
                if tok not in stop:
                    count += 1; pen_tokens.append(tok)
            # teacher-forced logits
            forced = [int(x) for x in torch.randint(0, 400, (n_forced,), generator=torch.Generator().manual_seed(77 + seed))]
            state, kvl = list(out[:2 * L]), kv_len
            for t in forced:
                c1, s1, kvl = rot_d(kvl)
                o = dec(*state, embed(torch.tensor([[t]], dtype=torch.int32)).float(), c1, s1, torch.zeros(1, 1, 1, 1, 1))
                state = list(o[:2 * L])
                logits.append(o[-1][0])
            logits = torch.stack(logits)
            n_head = len(pr_.head_ids) + len(query_ids) + len(pr_.suffix_ids)
            n_audio = n_prompt - n_head - len(pr_.tail_ids) - len(lang_tail)
            audio_hidden = prompt_embed[0, n_head:n_head + n_audio]
        # ---- oracle must reproduce the reference before the file is written ----
        fw = qo.fold_weights(raw, d)
        o_tok, st = qo.greedy_transcribe(pcm, fw, d, pr_, query_ids, lang_tail, max_new=max_new, return_stages=True)
        _, stf = qo.greedy_transcribe(pcm, fw, d, pr_, query_ids, lang_tail, forced=forced, return_stages=True)
        assert n_audio == qo.audio_token_count(n, d), (n_audio, qo.audio_token_count(n, d))
        for name, a, b in (("audio_hidden", st["audio_hidden"], audio_hidden), ("prompt_embed", st["prompt_embed"], prompt_embed[0]),
                           ("logits", stf["logits"], logits)):
            err = float((a - b).abs().max())
            print(f"case{case} {name}: oracle vs reference max|d| = {err:.3e}  (scale {float(b.abs().max()):.2f})")
            assert err <= 1e-3, name
        assert o_tok == tokens, (o_tok, tokens)
        o_pen = qo.greedy_transcribe(pcm, fw, d, pr_, query_ids, lang_tail, max_new=p_new, repeat_penalty=0.8, penalty_range=10)
        assert o_pen == pen_tokens, (o_pen, pen_tokens)
        np.savez_compressed(OUT / f"qwen_tiny_case{case}.npz", seed=seed, pcm=pcm, query_ids=np.array(query_ids, np.int32),
                            language_tail_ids=np.array(lang_tail, np.int32), n_prompt=n_prompt, n_audio=n_audio,
                            features=st["features"].numpy(), audio_hidden=audio_hidden.numpy(),
                            forced_tokens=np.array(forced, np.int32), forced_logits=logits.numpy(),
                            tokens=np.array(tokens, np.int32), max_new=max_new,
                            penalty_tokens=np.array(pen_tokens, np.int32), penalty_max_new=p_new)
        print(f"case{case}: prompt {n_prompt} ({n_audio} audio tokens), greedy {tokens}, penalty-greedy {pen_tokens}")


if __name__ == "__main__":
    main()

"""Mint golden vectors from the reference's own nn.Module wrappers (DEV CONTAINER ONLY).

Run:  python oracle/gen_golden.py            (needs /root/reference + transformers)
Out:  tests/golden/whisper_tiny_case{0,1,2}.npz

Each case = seeded synthetic PCM + seeded synthetic checkpoint pushed through
the AST-extracted reference modules (oracle/ref_loader.py) following the host
protocol of Whisper/Inference_Whisper_ONNX.py (probe -> prefill -> decode).
The fixtures are what pins ``oracle/whisper_oracle.py`` (tests/test_oracle_golden.py)
and, through it, the CUDA engine.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_loader, whisper_oracle as wo   # noqa: E402
from b200asr.synth import synth_pcm                    # noqa: E402

PROMPT = [3, 10, 11, 12]           # [SOT, lang, task, notimestamps] stand-ins in the 1000-token vocab
SOT = 3
EOS = 2
NO_SPEECH = 13
SUPPRESS = [1, 5, 7, 13] + list(range(900, 910))
BEGIN_SUPPRESS = [220, EOS]
LANG_IDS = list(range(20, 40))
N_STEPS = 6
FORCED = [17, 400, 23, 999, 0, 512]

CASES = [
    dict(name="whisper_tiny_case0", seed=0, clip=0, n_samples=32000),
    dict(name="whisper_tiny_case1", seed=1, clip=1, n_samples=24160),   # 151 frames -> odd -> T_enc 76
    dict(name="whisper_tiny_case2", seed=2, clip=2, n_samples=16000),
]


def run_reference(case):
    dims = wo.TINY_TEST
    raw = wo.make_raw_weights(dims, case["seed"])
    mods = ref_loader.build_reference_whisper(raw, dims, SUPPRESS, BEGIN_SUPPRESS)
    pcm = synth_pcm(case["clip"], case["n_samples"])
    audio = wo.prepare_audio(pcm)
    enc, dec = mods["encoder"], mods["decoder"]
    L = dims.dec_layers
    grabbed = {}
    hooks = [
        enc.encoder.layers[1].self_attn_layer_norm.register_forward_pre_hook(
            lambda m, a: grabbed.__setitem__("enc_layer0", a[0].detach().clone())),
        enc.encoder.layer_norm.register_forward_pre_hook(
            lambda m, a: grabbed.__setitem__("enc_pre_ln", a[0].detach().clone())),
        enc.encoder.layer_norm.register_forward_hook(
            lambda m, a, o: grabbed.__setitem__("enc_out", o.detach().clone())),
        enc.encoder.layers[0].self_attn_layer_norm.register_forward_pre_hook(
            lambda m, a: grabbed.__setitem__("stem", a[0].detach().clone())),
    ]
    out = {}
    with torch.no_grad():
        power = mods["stft"](audio)
        mel = torch.matmul(enc.fbank, power).clamp(min=1e-10).log10()      # Export_Whisper.py:425-427
        mel = torch.maximum(mel, mel.max() - 8.0)
        mel = (mel + 4.0) * 0.25
        cross = enc(audio)
        for h in hooks:
            h.remove()
        ck, cv = list(cross[:L]), list(cross[L:])
        k0 = [torch.zeros(1, dims.n_heads, dims.head_dim, 0) for _ in range(L)]
        v0 = [torch.zeros(1, dims.n_heads, 0, dims.head_dim) for _ in range(L)]

        def launch(ids, hist, sk, sv):
            n = len(ids)
            emb = mods["embed"](torch.tensor([ids], dtype=torch.int32))
            if n > 1 or hist == 0:
                pe, mask, kvlen = mods["prefill"](torch.tensor([n]), torch.tensor([hist]))
            else:
                pe, kvlen = mods["decode"](torch.tensor([hist]))
                mask = torch.zeros(1, 1, 1)
            r = dec(*sk, *sv, *ck, *cv, emb, pe, mask)
            return list(r[:L]), list(r[L:2 * L]), r[-1]

        # probe with [SOT] (language detection / no-speech source logits)
        _, _, probe_logits = launch([SOT], 0, k0, v0)
        nsd = mods["ns"]["NO_SPEECH_DETECTION"](NO_SPEECH, torch.tensor(SUPPRESS), dims.vocab)
        out["probe_logits"] = probe_logits[0].numpy()
        out["no_speech_prob"] = nsd(probe_logits).numpy()
        lang = np.asarray(LANG_IDS)
        out["detected_language"] = np.int64(lang[np.argmax(out["probe_logits"][lang])])

        # free-running greedy: prefill + N_STEPS decode launches
        sk, sv, logits = launch(PROMPT, 0, k0, v0)
        free_logits = [logits[0].numpy().copy()]
        tok = int(mods["argmax"](mods["begin"](logits))[0, 0])
        free_tokens = [tok]
        hist = len(PROMPT)
        for _ in range(N_STEPS):
            sk, sv, logits = launch([tok], hist, sk, sv)
            hist += 1
            free_logits.append(logits[0].numpy().copy())
            tok = int(mods["argmax"](logits)[0, 0])
            free_tokens.append(tok)
        out["self_k_last_layer"] = sk[L - 1][0].numpy()          # (H, dh, kv)
        out["self_v_layer0"] = sv[0][0].numpy()                  # (H, kv, dh)

        # teacher-forced: prefill then feed FORCED tokens
        sk, sv, logits = launch(PROMPT, 0, k0, v0)
        forced_logits = [logits[0].numpy().copy()]
        hist = len(PROMPT)
        for t in FORCED:
            sk, sv, logits = launch([t], hist, sk, sv)
            hist += 1
            forced_logits.append(logits[0].numpy().copy())

        # penalty-greedy (range 3, value 0.8) selected ids, Inference_Whisper_ONNX.py:607-652
        pr, pvv = 3, 0.8
        sk, sv, logits = launch(PROMPT, 0, k0, v0)
        sel, save = mods["greedy"](mods["begin"](logits), torch.zeros(1, 0, dtype=torch.int32))
        pen_tokens = [int(sel[0, 0])]
        hist = len(PROMPT)
        generated = 1
        for _ in range(N_STEPS):
            sk, sv, logits = launch([pen_tokens[-1]], hist, sk, sv)
            hist += 1
            value = pvv if generated >= pr else 1.0
            lg = mods["penalty"](logits, save, torch.tensor([value]), torch.tensor([pr]))
            sel, save = mods["greedy"](lg, save)
            pen_tokens.append(int(sel[0, 0]))
            generated += 1

    out.update(
        pcm=pcm, seed=np.int64(case["seed"]),
        power_sub=power[0, :, ::4].numpy(), mel=mel[0].numpy(),
        stem=grabbed["stem"][0].numpy(), enc_layer0=grabbed["enc_layer0"][0].numpy(),
        enc_pre_ln=grabbed["enc_pre_ln"][0].numpy(), enc_out=grabbed["enc_out"][0].numpy(),
        cross_k_layer0=ck[0].numpy(), cross_v_last=cv[L - 1].numpy(),
        free_logits=np.stack(free_logits), free_tokens=np.asarray(free_tokens, dtype=np.int32),
        forced_logits=np.stack(forced_logits), forced_tokens=np.asarray(FORCED, dtype=np.int32),
        penalty_tokens=np.asarray(pen_tokens, dtype=np.int32),
        prompt=np.asarray(PROMPT, dtype=np.int32), suppress=np.asarray(SUPPRESS, dtype=np.int32),
        begin_suppress=np.asarray(BEGIN_SUPPRESS, dtype=np.int32),
        lang_ids=np.asarray(LANG_IDS, dtype=np.int32),
        mel_fbank_ref=enc.fbank[0].numpy()[:, ::8],
        stft_kernel_ref=mods["stft"].stft_kernel[:, 0, :].numpy()[::16],
    )
    return out


def main():
    if not ref_loader.reference_available():
        raise SystemExit("/root/reference is not present: goldens can only be minted in the dev container")
    outdir = ROOT / "tests" / "golden"
    outdir.mkdir(parents=True, exist_ok=True)
    for case in CASES:
        out = run_reference(case)
        path = outdir / (case["name"] + ".npz")
        np.savez_compressed(path, **out)
        print(path, {k: getattr(v, "shape", None) for k, v in out.items() if hasattr(v, "shape") and v.ndim}, path.stat().st_size)


if __name__ == "__main__":
    main()

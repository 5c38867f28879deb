"""Mint SenseVoice golden vectors from the REFERENCE module itself (DEV CONTAINER ONLY).

/root/reference/SenseVoice/Export_SenseVoice.py cannot be imported (module-level code loads a FunASR checkpoint and
exports), and `funasr` is not installed: SENSE_VOICE is AST-extracted and handed a stub that exposes exactly the
attributes the wrapper reads (encoder.encoders0 / encoders / tp_encoders / after_norm / tp_norm, per layer
self_attn.{h, d_k, linear_q_k_v, linear_out, fsmn_block}, feed_forward.{w_1, w_2}, norm1/2, in_size, size; ctc.ctc_lo;
blank_id; embed), filled with the oracle's seeded synthetic checkpoint.  The export-time scaling of the exporter's
main block (:356-364: embed and CMVN scale x sqrt(d)) is applied the same way.  Outputs -> tests/golden/.
"""
import ast
import sys
from pathlib import Path

import numpy as np
import torch
import torchaudio.compliance.kaldi as kaldi

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import sensevoice_oracle as so  # noqa: E402

REF = Path("/root/reference/SenseVoice/Export_SenseVoice.py")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def load_reference_class():
    src = REF.read_text()
    want = {"SENSE_VOICE", "_SKIP_LAYER_NORMALIZATION", "_INTEGER_DIVIDE"}
    body = [n for n in ast.parse(src).body if isinstance(n, (ast.ClassDef, ast.FunctionDef)) and n.name in want]
    ns = dict(torch=torch, kaldi=kaldi, LANGUAGE_PROMPT_TOKEN_IDS=so.LANGUAGE_PROMPT_TOKEN_IDS)
    exec(compile(ast.Module(body=body, type_ignores=[]), "ref_sensevoice", "exec"), ns)
    return ns["SENSE_VOICE"]


class _Attn(torch.nn.Module):
    def __init__(self, din, d, H, k):
        super().__init__()
        self.h, self.d_k = H, d // H
        self.linear_q_k_v = torch.nn.Linear(din, 3 * d)
        self.linear_out = torch.nn.Linear(d, d)
        self.fsmn_block = torch.nn.Conv1d(d, d, k, stride=1, padding=0, groups=d, bias=False)


class _FF(torch.nn.Module):
    def __init__(self, d, f):
        super().__init__()
        self.w_1, self.w_2 = torch.nn.Linear(d, f), torch.nn.Linear(f, d)


class _Layer(torch.nn.Module):
    def __init__(self, din, dims):
        super().__init__()
        d = dims.d_model
        self.in_size, self.size = din, d
        self.self_attn = _Attn(din, d, dims.n_heads, dims.fsmn_kernel)
        self.feed_forward = _FF(d, dims.ffn)
        self.norm1 = torch.nn.LayerNorm(din, eps=dims.ln_eps)
        self.norm2 = torch.nn.LayerNorm(d, eps=dims.ln_eps)


class _Encoder(torch.nn.Module):
    def __init__(self, dims):
        super().__init__()
        d = dims.d_model
        self.encoders0 = torch.nn.ModuleList([_Layer(dims.feat, dims) for _ in range(dims.n_blocks0)])
        self.encoders = torch.nn.ModuleList([_Layer(d, dims) for _ in range(dims.n_blocks)])
        self.tp_encoders = torch.nn.ModuleList([_Layer(d, dims) for _ in range(dims.n_tp_blocks)])
        self.after_norm = torch.nn.LayerNorm(d, eps=dims.ln_eps)
        self.tp_norm = torch.nn.LayerNorm(d, eps=dims.ln_eps)

    def output_size(self):
        return self.after_norm.normalized_shape[0]


class _Stub(torch.nn.Module):
    def __init__(self, dims):
        super().__init__()
        self.encoder = _Encoder(dims)
        self.ctc = torch.nn.Module()
        self.ctc.ctc_lo = torch.nn.Linear(dims.d_model, dims.vocab)
        self.blank_id = dims.blank_id
        self.embed = torch.nn.Embedding(dims.n_embed, dims.feat)


def build_stub(raw, dims):
    m = _Stub(dims).eval()
    layers = list(m.encoder.encoders0) + list(m.encoder.encoders) + list(m.encoder.tp_encoders)
    with torch.no_grad():
        m.embed.weight.copy_(raw["embed"])
        for i, layer in enumerate(layers):
            p = f"blk{i}."
            layer.norm1.weight.copy_(raw[p + "norm1.g"]); layer.norm1.bias.copy_(raw[p + "norm1.b"])
            layer.norm2.weight.copy_(raw[p + "norm2.g"]); layer.norm2.bias.copy_(raw[p + "norm2.b"])
            layer.self_attn.linear_q_k_v.weight.copy_(raw[p + "qkv.w"]); layer.self_attn.linear_q_k_v.bias.copy_(raw[p + "qkv.b"])
            layer.self_attn.linear_out.weight.copy_(raw[p + "out.w"]); layer.self_attn.linear_out.bias.copy_(raw[p + "out.b"])
            layer.self_attn.fsmn_block.weight.copy_(raw[p + "fsmn.w"].unsqueeze(1))
            layer.feed_forward.w_1.weight.copy_(raw[p + "w1.w"]); layer.feed_forward.w_1.bias.copy_(raw[p + "w1.b"])
            layer.feed_forward.w_2.weight.copy_(raw[p + "w2.w"]); layer.feed_forward.w_2.bias.copy_(raw[p + "w2.b"])
        m.encoder.after_norm.weight.copy_(raw["after_norm.g"]); m.encoder.after_norm.bias.copy_(raw["after_norm.b"])
        m.encoder.tp_norm.weight.copy_(raw["tp_norm.g"]); m.encoder.tp_norm.bias.copy_(raw["tp_norm.b"])
        m.ctc.ctc_lo.weight.copy_(raw["ctc.w"]); m.ctc.ctc_lo.bias.copy_(raw["ctc.b"])
    return m


def synth_pcm(seed, n):
    g = torch.Generator().manual_seed(1234 + seed)
    x = torch.randn(n, generator=g) * 1638.0
    t = torch.arange(n, dtype=torch.float32) / 16000.0
    for f0 in (220.0, 440.0, 1760.0):
        x = x + 3000.0 * torch.sin(2 * torch.pi * f0 * t * (1.0 + 0.1 * seed))
    return x.round().clamp(-32768, 32767).to(torch.int16).numpy()


def main():
    SENSE_VOICE = load_reference_class()
    dims = so.TINY_TEST
    max_samples = 160000
    for case, (seed, n, lang) in enumerate([(0, 32000, 0), (1, 48160, 2), (2, 25999, 5)]):
        raw = so.make_raw_weights(dims, seed)
        stub = build_stub(raw, dims)
        scale = float(stub.encoder.output_size()) ** 0.5                       # Export_SenseVoice.py:361-364
        with torch.no_grad():
            stub.embed.weight.data *= scale
        cm = raw["cmvn_means"].reshape(1, 1, -1)
        cv = (raw["cmvn_vars"] * scale).reshape(1, 1, -1)
        sig = (max_samples - dims.win) // dims.hop + 1
        lfr_len = (sig + dims.lfr_n - 1) // dims.lfr_n
        with torch.no_grad():
            ref = SENSE_VOICE(stub, dims.d_model, dims.nfft, dims.win, dims.hop, sig, dims.n_mels, dims.sample_rate,
                              dims.pre_emphasis, dims.lfr_m, dims.lfr_n, lfr_len, cm, cv, True, False).eval()
            pcm = synth_pcm(seed, n)
            audio = torch.from_numpy(pcm.astype(np.float32)).reshape(1, 1, -1)
            tok, num = ref(audio, torch.tensor([lang], dtype=torch.int32))
            # intermediates through the reference's own sub-calls
            spectrum = torch.nn.functional.conv1d(audio, ref.fbank_kernel, stride=ref.hop_length)
            re, im = torch.split(spectrum * spectrum, ref.fbank_freq, dim=1)
            mel = torch.matmul((re + im).transpose(1, 2), ref.mel_filters).clamp(min=ref.log_eps).log()[0]
            frames = mel.shape[0]
            T = (frames + dims.lfr_n - 1) // dims.lfr_n
            idx = torch.minimum(ref.indices_mel[:T], torch.tensor(frames - 1))
            feats = mel[idx].reshape(-1, ref.feature_size)
            feats = (feats + ref.cmvn_means) * ref.cmvn_vars + ref.speech_position[:T]
            feats = torch.cat([ref.language_embed[torch.tensor([lang])], ref.system_embed, feats], dim=0)
            enc = ref.encode(feats)
            logits = ref.ctc_lo(enc)
        # oracle self-check against the reference module before anything is written
        fw = so.fold_weights(raw, dims, lfr_len)
        o_tok, st = so.transcribe(pcm, fw, dims, lang, return_stages=True)
        assert o_tok == tok.tolist(), (o_tok, tok.tolist())
        for name, a, b in (("mel", st["mel"], mel), ("feats", st["feats"], feats), ("enc_out", st["enc_out"], enc),
                           ("logits", st["logits"], logits)):
            err = float((a - b).abs().max())
            print(f"case{case} {name}: oracle vs reference max|d| = {err:.3e}")
            assert err <= 2e-3, name
        np.savez_compressed(OUT / f"sensevoice_tiny_case{case}.npz", seed=seed, pcm=pcm, language_idx=lang,
                            max_lfr=lfr_len, mel=mel.numpy(), feats=feats.numpy(), enc_out=enc.numpy(),
                            logits_sub=logits[:, :64].numpy(), frame_ids=logits.argmax(-1).numpy().astype(np.int32),
                            tokens=tok.numpy().astype(np.int32), num=num.numpy())
        print(f"case{case}: {int(num)} tokens from {T + 4} frames")


if __name__ == "__main__":
    main()

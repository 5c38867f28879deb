"""Mint golden vectors for the sampling head and the host-side helpers from the REFERENCE's own code
(DEV CONTAINER ONLY; /root/reference does not travel).  Output: tests/golden/whisper_heads.npz + host_logic.json.

* TOPK_TOPP_SAMPLING (Whisper/Export_Whisper.py:263-307) is AST-extracted and run under torch.manual_seed; the
  uniform noise it draws with torch.rand_like is recorded (same seed, same call) so the oracle / the CUDA head can
  be replayed with identical noise.
* prepare_audio_input / remove_repeated_parts (Whisper/Inference_Whisper_ONNX.py:103-139) and the window plan
  (:752-760) are AST-extracted from the driver script (it cannot be imported: it opens sessions at module level).
* ORT_IO.py is imported as is.
"""
import ast
import json
import sys
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def extract(path, names, ns):
    src = path.read_text()
    body = [n for n in ast.parse(src).body if isinstance(n, (ast.ClassDef, ast.FunctionDef)) and n.name in names]
    exec(compile(ast.Module(body=body, type_ignores=[]), str(path), "exec"), ns)
    return ns


def main():
    ns = extract(REF / "Whisper" / "Export_Whisper.py", {"TOPK_TOPP_SAMPLING"}, dict(torch=torch))
    head = ns["TOPK_TOPP_SAMPLING"]().eval()
    cases = []
    g = torch.Generator().manual_seed(7)
    for i, (vocab, k, p, t, rp, nprev) in enumerate([(1000, 10, 0.95, 0.8, 1.0, 0), (1000, 10, 0.95, 0.8, 1.3, 5),
                                                     (1000, 5, 0.5, 1.5, 1.1, 12), (51866, 10, 0.95, 0.8, 1.2, 30),
                                                     (1000, 1, 0.95, 0.8, 1.0, 3), (1000, 64, 0.3, 0.2, 2.0, 40)]):
        logits = torch.randn(1, vocab, generator=g) * 3.0
        prev = torch.randint(0, vocab, (1, nprev), generator=g, dtype=torch.int32)
        if nprev > 2:
            prev[0, 1] = prev[0, 0]                      # a duplicated id: scatter writes the same value twice
            prev[0, 2] = int(torch.argmax(logits))       # the arg-max itself gets penalised
        torch.manual_seed(100 + i)
        noise = torch.rand(1, k)                         # what rand_like will draw under the same seed
        torch.manual_seed(100 + i)
        with torch.no_grad():
            sid, save = head(logits, torch.tensor([t]), torch.tensor(k, dtype=torch.int64), torch.tensor([p]),
                             torch.tensor([rp]), prev.long() if False else prev.to(torch.int64))
        cases.append(dict(logits=logits.numpy(), prev=prev.numpy(), noise=noise.numpy(), k=k, p=p, t=t, rp=rp,
                          sampled=int(sid[0, 0]), save=save.numpy().astype(np.int32)))
    np.savez_compressed(OUT / "whisper_heads.npz", n=len(cases),
                        **{f"{key}_{i}": np.asarray(c[key]) for i, c in enumerate(cases) for key in c})

    # ---- host helpers ----
    ns2 = dict(np=np, USE_NORMALISE_AUDIO=False)
    extract(REF / "Whisper" / "Inference_Whisper_ONNX.py", {"prepare_audio_input", "remove_repeated_parts"}, ns2)
    rng = np.random.default_rng(3)
    pcm = (rng.standard_normal(4000) * 3000).clip(-32768, 32767).astype(np.int16).reshape(1, 1, -1)
    host = {"prepare": [], "repeat": [], "windows": []}
    for norm in (False, True):
        ns2["USE_NORMALISE_AUDIO"] = norm
        for dt in ("int16", "float32"):
            out = ns2["prepare_audio_input"](pcm, np.dtype(dt), audio_pcm_scale=32768)
            host["prepare"].append(dict(normalise=norm, dtype=dt, checksum=float(np.abs(out.astype(np.float64)).sum()),
                                        head=[float(x) for x in out.reshape(-1)[:8]]))
    for ids, thr in [([1, 2, 3, 4, 5, 6, 7], 3), ([5, 6, 7, 8, 9, 5, 6, 7, 8, 9, 1], 3), ([1, 2, 3, 1, 2, 3, 1, 2, 3, 4, 4], 3),
                     ([9, 9, 9, 9, 9, 9, 9, 9], 3), ([1, 2], 3)]:
        arr = np.asarray(ids)
        host["repeat"].append(dict(ids=ids, thr=thr, out=[int(x) for x in ns2["remove_repeated_parts"](arr, thr, arr.shape[-1])]))
    for audio_len, inp, slide in [(128000, 128000, 0), (300000, 128000, 0), (300000, 128000, 64000), (128001, 128000, 0),
                                  (480000, 160000, 80000), (1000, 128000, 0)]:
        stride = inp if slide <= 0 else slide                                   # Inference_Whisper_ONNX.py:752-760
        windows = 1 if audio_len <= inp else int(np.ceil((audio_len - inp) / stride)) + 1
        host["windows"].append(dict(audio_len=audio_len, input_len=inp, sliding=slide, windows=windows, stride=stride,
                                    aligned=(windows - 1) * stride + inp))
    np.save(OUT / "host_pcm.npy", pcm.reshape(-1))
    (OUT / "host_logic.json").write_text(json.dumps(host, indent=1))
    print("wrote", OUT / "whisper_heads.npz", OUT / "host_logic.json")


if __name__ == "__main__":
    main()

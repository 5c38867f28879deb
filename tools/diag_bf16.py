"""Stage-by-stage error report of the bf16 engine (tcgen05 and CUDA-core GEMMs) vs the fp32 oracle."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np
import torch
from gpu_common import GOLD, load_case, make_engine, maxdiff
from oracle import whisper_oracle as wo

g, raw, tensors = load_case(GOLD[0])
fw = wo.fold_weights(raw, wo.TINY_TEST, g["suppress"].tolist(), g["begin_suppress"].tolist())
with torch.no_grad():
    ck_o, cv_o, st = wo.encoder(wo.prepare_audio(g["pcm"]), fw, wo.TINY_TEST, keep_stages=True)
    sk, sv = wo.empty_self_kv(wo.TINY_TEST)
    sk, sv, lg_o = wo.decoder(torch.tensor([g["prompt"].tolist()], dtype=torch.int32), 0, sk, sv, ck_o, cv_o, fw, wo.TINY_TEST)
T = (len(g["pcm"]) // 160 + 1) // 2
for precision, tc in (("bf16", False), ("bf16", True)):
    eng = make_engine(tensors, precision, tc=tc)
    eng.encode(g["pcm"])
    ck = eng.get_stage("cross_k", 2 * 4 * T * 64).reshape(2, 4, T, 64)
    cv = eng.get_stage("cross_v", 2 * 4 * T * 64).reshape(2, 4, T, 64)
    for l in range(2):
        print(f"[{precision} tc={tc}] layer {l}: K err {maxdiff(ck[l].transpose(0, 2, 1), ck_o[l].numpy()):.3e} "
              f"V err {maxdiff(cv[l], cv_o[l].numpy()):.3e}   per-head V err",
              [round(maxdiff(cv[l][h], cv_o[l][h].numpy()), 3) for h in range(4)])
    eng.set_decode_options(stop_ids=[])
    logits, tok = eng.prefill(g["prompt"])
    k = eng.get_stage("self_k", 2 * 4 * 448 * 64).reshape(2, 4, -1, 64)
    v = eng.get_stage("self_v", 2 * 4 * 448 * 64).reshape(2, 4, -1, 64)
    for l in range(2):
        print(f"   self K[{l}] err {maxdiff(k[l].transpose(0, 2, 1), sk[l][0].numpy()):.3e}  self V[{l}] err {maxdiff(v[l], sv[l][0].numpy()):.3e}")
    print(f"   prefill logits err {maxdiff(logits[0], lg_o[0].numpy()):.3e}")
    eng.close()

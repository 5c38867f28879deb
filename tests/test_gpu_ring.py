"""Streaming decode kernel (decoder_ring.cu: TMA weight ring + flag-in-data exchanges) against the
barrier-based persistent kernel (decoder_mega.cu) and the fp32 goldens.  Same bf16 weights and the same
fp32 accumulation; only the summation order inside a dot product differs, so logits must agree to 5e-3
and free-running tokens must be identical wherever the golden margin is not a tie."""
import numpy as np
import pytest

from gpu_common import GOLD, load_case, make_engine, maxdiff
from b200asr.synth import synth_pcm

pytestmark = pytest.mark.gpu


def _forced(eng, pcm, prompt, forced):
    eng.encode(pcm)
    eng.set_decode_options(stop_ids=[])
    logits, tok = eng.prefill(prompt)
    out = [logits.copy()]
    for t in forced:
        logits, tok = eng.decode_step(token_in=np.full(eng.batch, t, np.int32))
        out.append(logits.copy())
    return np.stack(out, axis=1)


@pytest.mark.parametrize("path", GOLD, ids=[p.stem for p in GOLD])
def test_ring_vs_mega_logits_and_tokens(path):
    g, raw, tensors = load_case(path)
    res = {}
    for ring in (1, 0):
        eng = make_engine(tensors, "bf16")
        eng.set_option("ring", ring)
        lg = _forced(eng, g["pcm"], g["prompt"], g["forced_tokens"].tolist())
        eng.set_decode_options(stop_ids=[], generate_limit=12)
        toks = eng.transcribe(g["pcm"], g["prompt"], max_new=12)
        eng.set_decode_options(stop_ids=[], generate_limit=9, repeat_penalty=0.8, penalty_range=3)
        ptoks = eng.transcribe(g["pcm"], g["prompt"], max_new=9)
        # device loop after an explicit prefill
        eng.set_decode_options(stop_ids=[], generate_limit=7)
        eng.encode(g["pcm"])
        eng.prefill(g["prompt"], want_logits=False)
        loop = eng.decode()
        res[ring] = (lg, toks, ptoks, loop)
        eng.close()
    d = maxdiff(res[1][0], res[0][0])
    print("ring vs mega max |dlogit| =", d)
    assert d <= 5e-3
    assert maxdiff(res[1][0][0], g["forced_logits"]) <= 0.08
    assert res[1][1] == res[0][1]
    assert res[1][2] == res[0][2]
    assert res[1][3] == res[0][3]
    assert res[1][3][0] == res[1][1][0][:7]


def test_ring_stop_latch():
    g, raw, tensors = load_case(GOLD[0])
    eng = make_engine(tensors, "bf16")
    eng.set_decode_options(stop_ids=[], generate_limit=10)
    free = eng.transcribe(g["pcm"], g["prompt"], max_new=10)[0]
    stop = free[3]
    first = free.index(stop)
    eng.set_decode_options(stop_ids=[stop], generate_limit=10)
    got = eng.transcribe(g["pcm"], g["prompt"], max_new=10)[0]
    assert got == free[:first]
    eng.close()


@pytest.mark.parametrize("nb", [2, 3, 4])
def test_ring_batch(nb):
    g, raw, tensors = load_case(GOLD[1])
    n = 24160
    clips = np.stack([synth_pcm(20 + i, n) for i in range(nb)])
    forced = g["forced_tokens"].tolist()[:4]
    out = {}
    for ring in (1, 0):
        eng = make_engine(tensors, "bf16", max_batch=nb)
        eng.set_option("ring", ring)
        lg = _forced(eng, clips, g["prompt"], forced)
        eng.set_decode_options(stop_ids=[], generate_limit=8)
        toks = eng.transcribe(clips, g["prompt"], max_new=8)
        out[ring] = (lg, toks)
        eng.close()
    d = maxdiff(out[1][0], out[0][0])
    print(f"batch {nb}: ring vs mega max |dlogit| =", d)
    assert d <= 5e-3
    assert out[1][1] == out[0][1]

"""Edge cases of the Whisper path on the GPU: shortest and longest clips, the KV cache filled to max_target, batches
beyond the streaming kernel's 4 rows, stop tokens on the first step -- bf16 product path against the fp32 parity mode of
the same engine (itself held to the reference goldens in test_gpu_whisper_f32.py) and against the CPU oracle."""
import numpy as np
import pytest
import torch

from gpu_common import GOLD, load_case, make_engine, maxdiff
from oracle import whisper_oracle as wo
from b200asr.synth import synth_pcm

pytestmark = pytest.mark.gpu


def _fw(g):
    raw = wo.make_raw_weights(wo.TINY_TEST, int(g["seed"]))
    return wo.fold_weights(raw, wo.TINY_TEST, g["suppress"].tolist(), g["begin_suppress"].tolist())


@pytest.mark.parametrize("n_samples", [400, 479, 640, 1600])
def test_shortest_clips_match_oracle(n_samples):
    """n_fft samples is the shortest legal clip: T_mel = n // 160 (2 for 400), T_enc = (T_mel + 1) // 2 = 1."""
    g, raw, tensors = load_case(GOLD[0])
    pcm = synth_pcm(3, n_samples)
    with torch.no_grad():
        ref = wo.greedy_transcribe(pcm, _fw(g), wo.TINY_TEST, g["prompt"].tolist(), stop_tokens=[], max_new=5)
    eng = make_engine(tensors, "f32", max_samples=32000)
    eng.set_decode_options(stop_ids=[], generate_limit=5)
    eng.encode(pcm)
    logits, _ = eng.prefill(g["prompt"])
    assert maxdiff(logits[0], ref["step_logits"][0]) <= 1e-3
    assert eng.transcribe(pcm, g["prompt"], max_new=5)[0] == ref["tokens"]
    eng.close()
    with pytest.raises(Exception, match="n_samples out of range"):
        e2 = make_engine(tensors, "f32", max_samples=32000)
        try:
            e2.encode(pcm[:399])
        finally:
            e2.close()


def test_thirty_second_clip_beyond_fused_attention_limit():
    """T_enc = 1500 > 448: the encoder's attention runs the two-pass streaming tcgen05 kernel (round 1: unfused CUDA-core
    products), the streaming decoder walks twelve K/V boxes per head; bf16 against the fp32 parity mode of the same engine."""
    g, raw, tensors = load_case(GOLD[0])
    pcm = synth_pcm(5, 480000)
    out = {}
    for prec in ("f32", "bf16"):
        eng = make_engine(tensors, prec, max_samples=480000)
        eng.set_decode_options(stop_ids=[], generate_limit=0)
        eng.encode(pcm)
        logits, tok = eng.prefill(g["prompt"])
        lg = [logits[0].copy()]
        for t in g["forced_tokens"].tolist()[:4]:
            logits, _ = eng.decode_step(token_in=[t])
            lg.append(logits[0].copy())
        out[prec] = np.stack(lg)
        eng.close()
    d = maxdiff(out["bf16"], out["f32"])
    print("30 s clip: bf16 vs f32 max |dlogit| =", d)
    assert np.isfinite(out["bf16"]).all() and d <= 0.03


def test_cache_fills_to_max_target():
    """generate_limit = MAX_SEQ_LEN - prompt (Inference_Whisper_ONNX.py:821): 444 tokens, cache position 447 written, no more."""
    g, raw, tensors = load_case(GOLD[2])
    for prec in ("f32", "bf16"):
        eng = make_engine(tensors, prec)
        eng.set_decode_options(stop_ids=[], generate_limit=0)
        toks = eng.transcribe(g["pcm"], g["prompt"], max_new=0)[0]
        assert len(toks) == 448 - 4
        assert toks[:7] == g["free_tokens"].tolist() or prec == "bf16"
        # every utterance has latched its limit: further launches leave the loop at once and change nothing
        _, tok = eng.decode_step(want_logits=False)
        assert int(tok[0]) == toks[-1]
        eng.close()


def test_stop_token_on_first_step_yields_empty_transcript():
    g, raw, tensors = load_case(GOLD[0])
    first = int(g["free_tokens"][0])
    for prec in ("f32", "bf16"):
        eng = make_engine(tensors, prec)
        eng.set_decode_options(stop_ids=[first], generate_limit=9)
        assert eng.transcribe(g["pcm"], g["prompt"], max_new=9)[0] == []
        eng.close()


def test_batch_eight_uses_barrier_kernel_and_matches_singles():
    """Batches of 5-8 clips are beyond the streaming kernel's 4 rows and take the grid-barrier kernel."""
    g, raw, tensors = load_case(GOLD[1])
    clips = np.stack([synth_pcm(50 + i, 20000) for i in range(8)])
    eng = make_engine(tensors, "bf16", max_batch=8)
    eng.set_decode_options(stop_ids=[], generate_limit=6)
    tb = eng.transcribe(clips, g["prompt"], max_new=6)
    eng.encode(clips)
    lb, _ = eng.prefill(g["prompt"])
    singles = []
    for i in range(8):
        eng.encode(clips[i])
        l1, _ = eng.prefill(g["prompt"])
        singles.append(l1[0])
    assert maxdiff(lb, np.stack(singles)) <= 2e-2
    assert all(len(t) == 6 for t in tb)
    eng.close()

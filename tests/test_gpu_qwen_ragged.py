"""Ragged Qwen3-ASR batches (b200asr_qwen_*_ragged): clips of different lengths in one batch, every clip with the result it has
when it runs alone -- its own reflect padding and log-mel maximum, chunk / window key counts, audio rows, prompt length, RoPE
positions, cache length and generation_limit (the reference runs one clip per call, Inference_Qwen_ASR_ONNX.py:586-745).
Logits are compared bit for bit (fp32 and bf16): every kernel on the path works per row / per window / per (clip, head), and the
shorter clips' padding rows sit behind the causal mask."""
import numpy as np
import pytest

from oracle import qwen_oracle as qo
from b200asr import qwen as qw

pytestmark = pytest.mark.gpu
D = qw.QWEN_TINY_TEST
MAX_SAMPLES = 200000
# longest clip of the engine, one frame short of two chunks, one frame into the second window, the shortest clip the STFT takes
LENS = [200000, 31999, 128160, 480]
Q, L = (5, 6), (9,)


def _engine(seed, precision, max_batch=4):
    raw = qw.synth_qwen_checkpoint(D, seed)
    return qw.QwenEngine(D, qw.fold_qwen(raw, D), qw.TINY_PROMPT, precision=precision, max_batch=max_batch, max_samples=MAX_SAMPLES)


def _clips(lens=LENS, seed=11):
    rng = np.random.default_rng(seed)
    return [(rng.standard_normal(n) * 2500).clip(-32768, 32767).astype(np.int16) for n in lens]


def _forced(eng, pcm, forced, lens=None):
    n_prompt = eng.encode(pcm, Q, L, lens=lens)
    lg, _ = eng.prefill()
    out = [lg.copy()]
    for t in forced:
        lg, _ = eng.decode_step(token_in=np.full(eng.batch, t, np.int32))
        out.append(lg.copy())
    return n_prompt, np.stack(out, axis=1)        # [B, steps, vocab]


@pytest.mark.parametrize("precision", ["f32", "bf16"])
def test_qwen_ragged_batch_equals_single_clips(precision):
    clips = _clips()
    pcm, lens = qw.QwenEngine.pad_ragged(clips)
    pcm[1, lens[1]:] = 12345                       # whatever sits past a clip's end is never read
    eng = _engine(3, precision)
    forced = [11, 12, 13, 14]
    n_prompt, lb = _forced(eng, pcm, forced, lens=lens)
    assert n_prompt == [len(qw.TINY_PROMPT.head_ids) + len(Q) + len(qw.TINY_PROMPT.suffix_ids) + D.audio_tokens(int(n))
                        + len(qw.TINY_PROMPT.tail_ids) + len(L) for n in lens]
    singles = [_forced(eng, c, forced) for c in clips]
    assert [s[0] for s in singles] == n_prompt
    for b in range(len(clips)):
        d = float(np.abs(lb[b] - singles[b][1][0]).max())
        print(precision, "clip", b, "ragged vs alone max|dlogit| =", d)
        assert np.array_equal(lb[b], singles[b][1][0]), (b, d)
    # greedy streams: a fixed budget, then every clip to its own generation_limit (the shortest prompt generates the most)
    for mx in (6, -1):
        tb = eng.transcribe(pcm, Q, L, max_new=mx, lens=lens)
        ts = [eng.transcribe(c, Q, L, max_new=mx)[0] for c in clips]
        assert tb == ts, mx
    if precision == "f32":
        fw = qo.fold_weights(qo.make_raw_weights(qo.TINY_TEST, 3), qo.TINY_TEST)
        t12 = eng.transcribe(pcm, Q, L, max_new=12, lens=lens)
        for b in (1, 3):
            assert t12[b] == qo.greedy_transcribe(clips[b], fw, qo.TINY_TEST, qo.TINY_PROMPT, Q, L, max_new=12)
    eng.close()


def test_qwen_ragged_generation_limit_is_per_clip():
    clips = _clips()
    pcm, lens = qw.QwenEngine.pad_ragged(clips)
    eng = _engine(4, "f32")
    prompt = eng.encode(pcm, Q, L, lens=lens)
    raw = qw.synth_qwen_checkpoint(D, 4)
    noisy = qw.QwenEngine(D, qw.fold_qwen(raw, D), qw.QwenPrompt(qw.TINY_PROMPT.head_ids, qw.TINY_PROMPT.suffix_ids, qw.TINY_PROMPT.tail_ids, (D.vocab - 1,)),
                          precision="f32", max_batch=4, max_samples=MAX_SAMPLES)        # a stop id the streams rarely reach
    out = noisy.transcribe(pcm, Q, L, lens=lens)
    alone = [noisy.transcribe(c, Q, L)[0] for c in clips]
    assert out == alone                                  # (the long clips keep stepping after their limit without touching the others' cache)
    got = [len(t) for t in out]
    want = [D.max_seq_len - 10 - p for p in prompt]
    at_limit = [g for g, w in zip(got, want) if g == w]
    print("generated", got, "limits", want)
    assert len(set(at_limit)) >= 2 and all(g <= w for g, w in zip(got, want))
    eng.close(); noisy.close()


@pytest.mark.parametrize("opts", [{"graph": 0}, {"attn_split": 0}, {"attn_tc": 0}, {"attn_tiled": 0}, {"pdl": 0}], ids=lambda o: next(iter(o)))
def test_qwen_ragged_on_every_kernel_variant_bf16(opts):
    clips = _clips([90000, 40000, 128160], seed=12)
    pcm, lens = qw.QwenEngine.pad_ragged(clips)
    base = _engine(5, "bf16", max_batch=3)
    want = base.transcribe(pcm, Q, L, max_new=8, lens=lens)
    base.close()
    eng = _engine(5, "bf16", max_batch=3)
    for k, v in opts.items():
        eng.set_option(k, v)
    got = eng.transcribe(pcm, Q, L, max_new=8, lens=lens)
    alone = [eng.transcribe(c, Q, L, max_new=8)[0] for c in clips]
    assert got == alone, opts
    if "attn" not in next(iter(opts)):              # (other attention kernels sum in another order: streams may differ legitimately)
        assert got == want, opts
    eng.close()


def test_qwen_ragged_penalty_greedy_and_uniform_batches_interleaved():
    clips = _clips([70000, 31999, 70000], seed=13)
    pcm, lens = qw.QwenEngine.pad_ragged(clips)
    eng = _engine(6, "bf16", max_batch=3)
    eng.set_decode_options(0.8, 10)
    r1 = eng.transcribe(pcm, Q, L, max_new=12, lens=lens)
    u1 = eng.transcribe(pcm, Q, L, max_new=12)                   # the same buffer as a uniform batch: clip 1 now includes its padding
    r2 = eng.transcribe(pcm, Q, L, max_new=12, lens=lens)
    u2 = eng.transcribe(pcm, Q, L, max_new=12)
    assert r1 == r2 and u1 == u2
    assert r1[0] == u1[0] and r1[2] == u1[2]
    assert r1 == [eng.transcribe(c, Q, L, max_new=12)[0] for c in clips]
    # resident split (upload, then transcribe) takes the same path
    eng.upload(pcm, lens=lens)
    assert eng.transcribe_resident(Q, L, max_new=12) == r1
    eng.close()


def test_qwen_ragged_float_pcm_and_argument_checks():
    clips = _clips([50000, 20000], seed=14)
    pcm, lens = qw.QwenEngine.pad_ragged(clips)
    eng = _engine(7, "f32", max_batch=2)
    want = eng.transcribe(pcm, Q, L, max_new=5, lens=lens)
    assert eng.transcribe(pcm.astype(np.float32) / 32768.0, Q, L, max_new=5, lens=lens) == want
    with pytest.raises(RuntimeError, match="clip length out of range"):
        eng.transcribe(pcm, Q, L, lens=[50000, D.nfft - 1])
    with pytest.raises(RuntimeError, match="clip length out of range"):
        eng.transcribe(pcm, Q, L, lens=[50001, 20000])
    with pytest.raises(RuntimeError, match="longest clip"):
        eng.transcribe(pcm, Q, L, lens=[40000, 20000])
    with pytest.raises(ValueError):
        eng.transcribe(pcm, Q, L, lens=[50000])
    assert eng.transcribe(pcm, Q, L, max_new=5, lens=lens) == want            # the engine is usable after the refusals
    eng.close()

"""Host logic of WhisperPipeline on a CPU stand-in engine (oracle-backed, tests only): the ragged-batch protocol
(`transcribe_batch`) and the batched long-form windows (`BATCH_WINDOWS`) must give exactly what the sequential per-clip loop of
the reference script (Inference_Whisper_ONNX.py:766-827, restated in `transcribe_pcm`) gives.  The same statements are held
on the CUDA engine by tests/test_gpu_ragged.py and tests/test_gpu_protocol.py."""
import json

import numpy as np

from oracle import whisper_oracle as wo
from oracle.cpu_engine import OracleWhisperEngine
from b200asr.cli import whisper_metadata
from b200asr.config import WHISPER_TINY_TEST as DIMS
from b200asr.synth import synth_pcm
from b200asr.whisper_infer import InferenceOptions, WhisperPipeline

SUP, BEG = [1, 5, 7, 13], [220, 2]
GEN = {"lang_to_id": {f"<|l{t}|>": t for t in range(20, 40)} | {"<|en|>": 10}, "task_to_id": {"transcribe": 11},
       "no_timestamps_token_id": 12, "decoder_start_token_id": 3, "eos_token_id": 2, "no_speech_token_id": 13}


class BatchedOracleEngine:
    """`max_batch` single-clip oracle engines behind the batch surface WhisperPipeline uses (encode with `lens`, per-clip prompts,
    the device-resident greedy loop's stop / limit rules)."""

    def __init__(self, fw, max_batch):
        self.dims, self.max_batch = DIMS, max_batch
        self.engines = [OracleWhisperEngine(DIMS, fw, SUP) for _ in range(max_batch)]
        self.batch, self.n_prompt, self.first = 0, 0, []
        self.opts = {}

    def set_decode_options(self, **kw):
        self.opts = kw
        for e in self.engines:
            e.set_decode_options(**kw)

    def set_sampling(self, **kw):
        pass

    def encode(self, pcm, lens=None):
        pcm = np.asarray(pcm)
        pcm = pcm.reshape(1, -1) if pcm.ndim == 1 else pcm.reshape(pcm.shape[0], pcm.shape[-1])
        self.batch = pcm.shape[0]
        for b in range(self.batch):
            self.engines[b].encode(pcm[b, :(pcm.shape[1] if lens is None else int(lens[b]))])

    def prefill(self, prompt, want_logits=True):
        p = np.asarray(prompt, np.int32)
        p = np.tile(p[None], (self.batch, 1)) if p.ndim == 1 else p
        self.n_prompt = p.shape[1]
        outs = [self.engines[b].prefill(p[b]) for b in range(self.batch)]
        self.first = [int(t[0]) for _, t in outs]
        return np.concatenate([l for l, _ in outs], axis=0), np.asarray(self.first, np.int32)

    def no_speech_prob(self, token):
        return np.concatenate([self.engines[b].no_speech_prob(token) for b in range(self.batch)])

    def decode(self):
        stop = set(int(s) for s in self.opts.get("stop_ids", ()))
        limit = DIMS.max_target - self.n_prompt
        cfg = int(self.opts.get("generate_limit", 0))
        if cfg > 0:
            limit = min(limit, cfg)
        res = []
        for b in range(self.batch):
            toks, sel = [], self.first[b]
            if sel not in stop and limit > 0:
                toks.append(sel)
                while len(toks) < limit:
                    _, s = self.engines[b].decode_step(want_logits=False)
                    if int(s[0]) in stop:
                        break
                    toks.append(int(s[0]))
            res.append(toks)
        return res


def _pipe(max_batch, **opt):
    raw = wo.make_raw_weights(wo.TINY_TEST, 4, pos_scale=100.0)
    fw = wo.fold_weights(raw, wo.TINY_TEST, SUP, BEG)
    eng = BatchedOracleEngine(fw, max_batch)
    md = whisper_metadata(DIMS, GEN)
    pipe = WhisperPipeline(eng, md, InferenceOptions(**opt))
    pipe.max_seq_len = DIMS.max_target
    return pipe


def test_transcribe_batch_equals_clip_by_clip_on_the_oracle():
    pipe = _pipe(3, REPEAT_PENALTY=0.8, PENALTY_RANGE=3, NO_SPEECH_THRESHOLD=2.0)
    pipe.engine.set_decode_options(stop_ids=[2], generate_limit=0)
    clips = [synth_pcm(50 + i, n) for i, n in enumerate((16000, 9600, 12800))]
    pipe.max_seq_len = DIMS.max_target
    batch = pipe.transcribe_batch(clips)
    for b, clip in enumerate(clips):
        one = pipe.transcribe_pcm(clip)
        assert batch[b].tokens == one.tokens and batch[b].language_token == one.language_token
        assert abs(batch[b].no_speech_probability - one.no_speech_probability) < 1e-6
    assert any(len(r.tokens) > 3 for r in batch)
    # a clip classified as silence keeps an empty token list, the others are unaffected
    quiet = _pipe(3, NO_SPEECH_THRESHOLD=0.0)
    assert all(r.no_speech and r.tokens == [] for r in quiet.transcribe_batch(clips))


def test_batched_windows_equal_the_sequential_loop_on_the_oracle():
    pcm = synth_pcm(9, 30000)
    out = {}
    for batched in (True, False):
        pipe = _pipe(4, REPEAT_PENALTY=1.0, NO_SPEECH_THRESHOLD=2.0, INPUT_AUDIO_LENGTH=6400, SLIDING_WINDOW=5600, BATCH_WINDOWS=batched)
        r = pipe.transcribe_pcm(pcm)
        out[batched] = (r.tokens, r.windows, r.language_token, r.decode_steps)
    assert out[True] == out[False]
    assert out[True][1] == 6 and len(out[True][0]) > 6
    assert json.dumps(out[True][0])          # plain ints
